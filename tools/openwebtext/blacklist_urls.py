"""Filter a directory of url lists: drop blacklisted domains / file extensions, malformed, too-short and duplicate urls
(parity: tools/openwebtext/blacklist_urls.py).   usage: blacklist_urls.py <dir with url files> <clean url file>"""
import glob
import re
import sys
import time

from textutils import registered_domain

# media / social / shopping / link-shortener / adult domains whose pages are not prose
DOMAIN_BLACKLIST = frozenset("""
500px aapks akamaihd amazon apple artifactfire artstation awwni bandcamp battleforthenet coinscalendar dailymotion
deviantart discord discordapp dlapkandroid dropbox e621 ebay edealinfo erome eroshare explosm facebook fbcdn flickr
furaffinity futhead gatopardo gfycat gifsound gifsoup giphy github google gunprime gyazo hotdealstar imagefap imageshack
imgflip imgur instagram karmadecay kryptocal kym-cdn liveleak livememe lmgtfy magaimg memegenerator minorplanetcenter
minus mobafire morejpeg nocookie pcpartpicker photobucket pinimg pinterest pixiv pornhub prntscr puu qkme quickmeme
radd redd reddit reddit-stream redditlog redditmedia reddituploads redtube reupp reverb roanoke rollingstone sli soundcloud
soundgasm spankbang spotify strawpoll streamable timeanddate tinypic touhouradio tumblr twimg twitch twitter vid vimeo
vine vkaao vocaroo voyagefusion walmart wciu wikimedia wikipedia xhamster xkcd xvideos youtu youtube youtubedoubler ytimg
zillexplorer""".split())
EXTENSION_BLACKLIST = (".3gp", ".7z", ".ai", ".aif", ".apk", ".app", ".avi", ".bin", ".bmp", ".bz2", ".css", ".csv",
                       ".dat", ".deb", ".dmg", ".doc", ".docx", ".exe", ".gif", ".gifv", ".gz", ".iso", ".jar",
                       ".jpeg", ".jpg", ".js", ".log", ".mid", ".midi", ".mkv", ".mov", ".mp3", ".mp4", ".mpeg",
                       ".mpg", ".ogg", ".ogv", ".otf", ".pdf", ".pkg", ".png", ".pps", ".ppt", ".pptx", ".psd", ".py",
                       ".qt", ".ram", ".rar", ".sql", ".svg", ".swf", ".tar.gz", ".tar", ".tgz", ".tiff", ".ttf",
                       ".txt", ".wav", ".webm", ".wma", ".wmv", ".xls", ".xlsx", ".xml", ".xz", ".zip")
_URL = re.compile(r"^(?:http)s?://(?:(?:[A-Z0-9](?:[A-Z0-9-]{0,61}[A-Z0-9])?\.)+(?:[A-Z]{2,6}\.?|[A-Z0-9-]{2,}\.?)|"
                  r"localhost|\d{1,3}\.\d{1,3}\.\d{1,3}\.\d{1,3})(?::\d+)?(?:/?|[/?]\S+)$", re.IGNORECASE)


def domain_is_in_blacklist(url):
    return registered_domain(url)[0] in DOMAIN_BLACKLIST


def extention_is_in_blacklist(url):
    return url.split("?")[0].lower().endswith(EXTENSION_BLACKLIST)


def url_is_malformed(url):
    return _URL.match(url) is None


if __name__ == "__main__":
    path, output = sys.argv[1], sys.argv[2]
    files = sorted(glob.glob(path + "/*.txt"))
    print("> found {} files".format(len(files)))
    urls, counts, t0 = set(), dict(domain=0, extension=0, short=0, malformed=0, duplicate=0), time.time()
    n = 0
    for filename in files:
        with open(filename, "r") as f:
            for line in f:
                url = line.strip()
                n += 1
                reason = ("domain" if domain_is_in_blacklist(url) else "extension" if extention_is_in_blacklist(url)
                          else "short" if len(url) <= 8 else "malformed" if url_is_malformed(url)
                          else "duplicate" if url in urls else None)
                if reason:
                    counts[reason] += 1
                else:
                    urls.add(url)
                if n % 100000 == 0:
                    print("[PROGRESS] {:.1f}s urls {} kept {} dropped {}".format(time.time() - t0, n, len(urls), counts),
                          flush=True)
    print("[FINAL] urls {} kept {} dropped {}".format(n, len(urls), counts), flush=True)
    with open(output, "w") as f:
        for url in urls:
            f.write(url + "\n")
    print("done :-)")
